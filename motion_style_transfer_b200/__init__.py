"""motion_style_transfer_b200 -- B200-native Y-Net forecasting hot path (MoSA).

Drop-in for the reference's ``models/ynet.py::YNet`` / ``models/trainer.py::YNetTrainer`` path:
same module tree and state-dict keys, same function signatures in ``utils``; every computation
runs in hand-written sm_100a CUDA behind the C ABI in ``include/ynet_b200.h``.  No CPU fallback.
"""
__version__ = '0.1.0'
