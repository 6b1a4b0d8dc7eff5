"""Constants of remove-background used on the hot path (values of the reference's
cellbender/remove_background/consts.py, cited by line)."""
RANDOM_SEED = 1234  # consts.py:4
CELL_PROB_CUTOFF = 0.5  # consts.py:21
TRAINING_FRACTION = 0.9  # consts.py:31
DEFAULT_BATCH_SIZE = 128  # consts.py:34
FRACTION_EMPTIES = 0.2  # consts.py:37
RHO_ALPHA_PRIOR = 1.5  # consts.py:40
RHO_BETA_PRIOR = 50.0  # consts.py:41
RHO_PARAM_MIN = 0.001  # consts.py:44
RHO_PARAM_MAX = 1000.0  # consts.py:45
EPSILON_PRIOR = 50.0  # consts.py:48
PHI_LOC_PRIOR = 0.2  # consts.py:51
PHI_SCALE_PRIOR = 0.2  # consts.py:52
D_CELL_SCALE_INIT = 0.02  # consts.py:60
P_LOGIT_SCALE = 2.0  # consts.py:63
NBPC_MU_EPS_SAFEGAURD = 1e-10  # consts.py:72
NBPC_ALPHA_EPS_SAFEGAURD = 1e-10  # consts.py:73
NBPC_LAM_EPS_SAFEGAURD = 1e-10  # consts.py:74
POISSON_EPS_SAFEGAURD = 1e-10  # consts.py:75
REG_SCALE_AMBIENT_EXPRESSION = 0.01  # consts.py:78
REG_SCALE_SOFT_SUPERVISION = 10.0  # consts.py:79
REG_LOGIT_MEAN = 10.0  # consts.py:82
REG_LOGIT_SCALE = 0.2  # consts.py:83
REG_LOGIT_SOFT_SCALE = 1.0  # consts.py:84
PROB_POSTERIOR_BATCH_SIZE = 128  # consts.py:103
MAX_BATCH_SIZE = 256  # consts.py:112
SMALLEST_ALLOWED_BATCH = 4  # consts.py:113

USE_EXACT_LOG_PROB = False  # consts.py:68
# max_poisson of NegativeBinomialPoissonConv when USE_EXACT_LOG_PROB is set (reference model.py:355)
NBPC_EXACT_MAX_POISSON = 100
