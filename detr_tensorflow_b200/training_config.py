"""Mirror of detr_tf/training_config.py: same flags, same attribute names.  Learning rates are plain floats
(the reference wraps them in tf.Variable only so they can change under a traced graph, :66-68; here they live in a
small device array that the optimizer kernel reads, so a captured CUDA graph sees updates too)."""
import argparse
import os


def training_config_parser():
    """training_config.py:6-38 (the reference declares the lr flags as type=bool, :31-33 -- a bug; floats here)."""
    parser = argparse.ArgumentParser()
    parser.add_argument("--data_dir", type=str, required=False, help="Path to the dataset directory")
    parser.add_argument("--img_dir", type=str, required=False, help="Image directory relative to data_dir")
    parser.add_argument("--ann_file", type=str, required=False, help="Annotation file relative to data_dir")
    parser.add_argument("--ann_dir", type=str, required=False, help="Annotation directory relative to data_dir")
    parser.add_argument("--background_class", type=int, required=False, default=0, help="Default background class")
    parser.add_argument("--train_backbone", action="store_true", required=False, default=False, help="Train backbone")
    parser.add_argument("--train_transformers", action="store_true", required=False, default=False, help="Train transformers")
    parser.add_argument("--train_nlayers", action="store_true", required=False, default=False, help="Train new layers")
    parser.add_argument("--finetuning", default=False, required=False, action="store_true", help="Load the model weight before to train")
    parser.add_argument("--batch_size", type=int, required=False, default=1, help="Batch size to use to train the model")
    parser.add_argument("--gradient_norm_clipping", type=float, required=False, default=0.1, help="Gradient norm clipping")
    parser.add_argument("--target_batch", type=int, required=False, default=None,
                        help="When running on a single GPU, aggretate the gradient before to apply.")
    parser.add_argument("--backbone_lr", type=float, required=False, default=1e-5, help="Backbone learning rate")
    parser.add_argument("--transformers_lr", type=float, required=False, default=1e-4, help="Transformers learning rate")
    parser.add_argument("--nlayers_lr", type=float, required=False, default=1e-4, help="New layers learning rate")
    parser.add_argument("--log", required=False, action="store_true", default=False, help="Log into wandb")
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
