"""Import-only stand-in for the MSN `emd` extension (evaluation metric, off the train step)."""


def forward(*_a, **_k):
    raise NotImplementedError("emd is not part of the TPU-GAN train step; not provided by tpugan_b200")


backward = forward
