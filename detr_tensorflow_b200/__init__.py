"""detr_tensorflow_b200 -- B200-native (sm_100a) drop-in for the DETR train-step hot path of
Visual-Behavior/detr-tensorflow: get_detr_model / get_losses / hungarian_matching / setup_optimizers / training.fit.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic runs in hand-written CUDA
behind the C ABI of include/detrb.h (libdetrb.so).  There is no CPU fallback."""
from .training_config import TrainingConfig, DataConfig, training_config_parser  # noqa: F401
from .networks.detr import get_detr_model  # noqa: F401
from .loss.loss import get_losses  # noqa: F401
from .loss.hungarian_matching import hungarian_matching  # noqa: F401
from .optimizers import setup_optimizers  # noqa: F401
from . import training  # noqa: F401
from . import inference  # noqa: F401
from .inference import get_model_inference  # noqa: F401
from . import data  # noqa: F401
