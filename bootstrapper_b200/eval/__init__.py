"""Consumers of the segmentation that run on the same device arrays."""
from .aff_errors import add_aff_errors  # noqa: F401
