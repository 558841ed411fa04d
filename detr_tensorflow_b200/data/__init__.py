from .processing import normalized_images, pad_labels, normalisation_lut  # noqa: F401
