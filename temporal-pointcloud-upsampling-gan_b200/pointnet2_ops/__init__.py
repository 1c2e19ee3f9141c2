"""Drop-in for ``pointnet2_ops`` (erikwijmans/Pointnet2_PyTorch pointnet2_ops_lib) as
imported by the reference (gcn_lib/pointnet/gcn.py:9, discriminator.py:7-8)."""
from . import pointnet2_utils  # noqa: F401

__version__ = "3.0.0+tpugan_b200"
