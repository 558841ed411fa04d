"""Mirror of detr_tf/training_config.py: same flags, same attribute names.  Learning rates are plain floats
(the reference wraps them in tf.Variable only so they can change under a traced graph, :66-68; here they live in a
small device array that the optimizer kernel reads, so a captured CUDA graph sees updates too)."""
import argparse
import os


# flag -> (kind, default, help); kinds: "path" (optional str), "int", "float", "switch" (store_true)
_FLAGS = (
    ("data_dir", "path", None, "dataset directory"),
    ("img_dir", "path", None, "image directory, relative to data_dir"),
    ("ann_file", "path", None, "annotation file, relative to data_dir"),
    ("ann_dir", "path", None, "annotation directory, relative to data_dir"),
    ("background_class", "int", 0, "index of the background class"),
    ("train_backbone", "switch", False, "update the backbone group"),
    ("train_transformers", "switch", False, "update the transformers group"),
    ("train_nlayers", "switch", False, "update the fine-tuning layers"),
    ("finetuning", "switch", False, "load model weights before training"),
    ("batch_size", "int", 1, "images per step"),
    ("gradient_norm_clipping", "float", 0.1, "per-variable clipnorm"),
    ("target_batch", "int", None, "accumulate gradients until target_batch images were seen"),
    ("backbone_lr", "float", 1e-5, "learning rate of the backbone group"),
    ("transformers_lr", "float", 1e-4, "learning rate of the transformers group"),
    ("nlayers_lr", "float", 1e-4, "learning rate of the fine-tuning layers"),
    ("log", "switch", False, "log to wandb (hook only)"),
)


def training_config_parser():
    """The command-line flags of training_config.py:6-38, same names and defaults (the reference declares the three lr flags
    as type=bool, :31-33 -- a bug; floats here)."""
    parser = argparse.ArgumentParser()
    for flag, kind, default, text in _FLAGS:
        if kind == "switch":
            parser.add_argument("--" + flag, action="store_true", default=default, help=text)
        else:
            parser.add_argument("--" + flag, type={"path": str, "int": int, "float": float}[kind], default=default, help=text)
    return parser


class TrainingConfig:
    """training_config.py:41-103"""

    def __init__(self):
        self.data_dir, self.img_dir, self.ann_dir, self.ann_file = None, None, None, None
        self.data = DataConfig(data_dir=None, img_dir=None, ann_file=None, ann_dir=None)
        self.background_class = 0
        self.image_size = 376, 672
        self.train_backbone = False
        self.train_transformers = False
        self.train_nlayers = False
        self.finetuning = False
        self.batch_size = 1
        self.gradient_norm_clipping = 0.1
        self.target_batch = 1
        self.backbone_lr = 1e-5
        self.transformers_lr = 1e-4
        self.nlayers_lr = 1e-4
        self.nlayers = []
        self.global_step = 0
        self.log = False
        self.normalized_method = "torch_resnet"

    def add_nlayers(self, layers):
        self.nlayers = [getattr(l, "name", l) for l in layers]

    def update_from_args(self, args):
        args = vars(args)
        for key in args:
            setattr(self, key, args[key])
        self.data = DataConfig(data_dir=self.data_dir, img_dir=self.img_dir, ann_file=self.ann_file, ann_dir=self.ann_dir)


class DataConfig:
    """training_config.py:106-112"""

    def __init__(self, data_dir=None, img_dir=None, ann_file=None, ann_dir=None):
        self.data_dir = data_dir
        self.img_dir = os.path.join(data_dir, img_dir) if data_dir is not None and img_dir is not None else None
        self.ann_file = os.path.join(self.data_dir, ann_file) if ann_file is not None else None
        self.ann_dir = os.path.join(self.data_dir, ann_dir) if ann_dir is not None else None
