from .processing import normalized_images, pad_labels, normalisation_lut  # noqa: F401
from .transformation import detr_transform, detr_transform_batch, sample_geometry, transform_boxes, resample_batch  # noqa: F401
