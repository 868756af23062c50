for mode in fused3 modular4; do for lanes in 1 2; do
echo "== $mode lanes=$lanes"; MMF_BENCH_STEP=$mode MMF_BENCH_QUICK=1 MMF_BENCH_INFLIGHT=$lanes python bench.py --steps 64 --warmup 8 2>&1 | tail -n 1
done; done
