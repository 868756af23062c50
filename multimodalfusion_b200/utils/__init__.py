from .loss_utils import (CoxSurvLoss, CrossEntropySurvLoss, NLLSurvLoss, RankingNLLSurvLoss,  # noqa: F401
                         RankingSurvLoss, ce_loss, nll_loss, ranking_loss)
from .utils import (dfs_freeze, dfs_unfreeze, init_max_weights, initialize_weights, l1_reg_all,  # noqa: F401
                    l1_reg_modules)
from .optim import FusedAdam, concordance_index, get_optim, to_percentiles  # noqa: F401
