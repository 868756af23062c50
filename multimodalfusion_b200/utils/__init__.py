from .loss_utils import (CoxSurvLoss, CrossEntropySurvLoss, NLLSurvLoss, RankingNLLSurvLoss,  # noqa: F401
                         RankingSurvLoss, ce_loss, nll_loss, ranking_loss)
from .utils import init_max_weights, initialize_weights  # noqa: F401
from .optim import FusedAdam, concordance_index, get_optim, to_percentiles  # noqa: F401
