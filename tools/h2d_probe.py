"""How fast can a 32 MiB bf16 bag reach the device? (the e2e number of bench.py is bound by this copy)
variants: one cudaMemcpyAsync from pinned memory; the bag split over 2 / 4 streams; write-combined pinned memory;
a copy kernel reading mapped (zero-copy) pinned memory."""
import ctypes, os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
dev = torch.device("cuda")
N = 16384
nbytes = N * 2048
src = torch.empty(N, 1024, dtype=torch.bfloat16).pin_memory()
src.copy_((0.5 * torch.randn(N, 1024).abs()).to(torch.bfloat16))
dst = torch.empty(N, 1024, dtype=torch.bfloat16, device=dev)


def timeit(fn, reps=12):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def report(name, ms):
    print(f"{name:44s} {ms * 1e3:8.0f} us  {nbytes / ms / 1e6:6.1f} GB/s  -> {N / ms / 1e3:5.1f} M patches/s ceiling")


report("1 x cudaMemcpyAsync (pinned)", timeit(lambda: dst.copy_(src, non_blocking=True)))
for k in (2, 4):
    streams = [torch.cuda.Stream() for _ in range(k)]
    rows = N // k

    def split():
        cur = torch.cuda.current_stream()
        evs = []
        for i, s in enumerate(streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                dst[i * rows:(i + 1) * rows].copy_(src[i * rows:(i + 1) * rows], non_blocking=True)
            cur.wait_stream(s)
    report(f"{k} streams x 1/{k} of the bag", timeit(split))
# write-combined pinned allocation through the runtime
rt = ctypes.CDLL("libcudart.so")
ptr = ctypes.c_void_p()
rc = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(4))   # cudaHostAllocWriteCombined
if rc == 0:
    buf = (ctypes.c_char * nbytes).from_address(ptr.value)
    wc = torch.frombuffer(buf, dtype=torch.bfloat16).view(N, 1024)
    wc.copy_(src)
    report("1 x cudaMemcpyAsync (write-combined pinned)", timeit(lambda: rt.cudaMemcpyAsync(
        ctypes.c_void_p(dst.data_ptr()), ptr, ctypes.c_size_t(nbytes), 1, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))))
# zero-copy: device kernel reads mapped pinned memory (torch copy kernel on a host-mapped tensor)
dptr = ctypes.c_void_p()
if rt.cudaHostGetDevicePointer(ctypes.byref(dptr), ctypes.c_void_p(src.data_ptr()), 0) == 0:
    class _Arr:  # __cuda_array_interface__ wrapper of the mapped pointer
        pass
    a = _Arr()
    a.__cuda_array_interface__ = {"shape": (N * 1024,), "typestr": "<i2", "data": (dptr.value, False), "version": 2}
    mapped = torch.as_tensor(a, device=dev).view(torch.bfloat16).view(N, 1024)
    report("copy kernel from mapped pinned memory", timeit(lambda: dst.copy_(mapped)))
