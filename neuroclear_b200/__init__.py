"""neuroclear_b200 — B200-native (sm_100a) hot path of peterhpark/neuroclear.

Host-side mirror of the reference's interface for the diced ``unet_deconv`` inference path
(``networks.define_G``, ``DiceImageDataSet``, ``Assemble_Dice``, ``Volume``) over the C ABI in
``include/neuroclear_b200.h``.  PyTorch provides device memory, streams and ``torch.distributed``; every kernel on
the path is hand-written CUDA in ``csrc/``.
"""
__version__ = "0.1.0"
