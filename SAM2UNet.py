"""Drop-in module name of the reference (`from SAM2UNet import SAM2UNet`, /root/reference/train.py:18, test.py:11)."""
from sam2_unet_b200 import SAM2UNet  # noqa: F401

__all__ = ["SAM2UNet"]
