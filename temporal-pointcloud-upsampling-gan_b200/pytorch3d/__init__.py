"""Drop-in for the slice of pytorch3d that TPU-GAN imports (``pytorch3d.ops``), backed by
libtpugan_b200.so.  Not the real pytorch3d."""
__version__ = "0.0.0+tpugan_b200"
