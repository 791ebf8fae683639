"""sqair_b200 -- B200-native implementation of SQAIR's per-frame Discover/Propagate hot path.

Host side: the reference's operator surface (`model.Model`, `seq.SequentialAIR`,
`sqair_modules.{Discover,Propagate,SQAIRTimestep}`, `core`, `propagate`, `modules`, `targets`,
`tf_flags`, `common_model_flags`, `configs/mlp_mnist_model.py`) as light specification objects.
Compute: hand-written sm_100a CUDA behind the C ABI of include/sqair_b200.h (csrc/), loaded by
`_capi`; there is no CPU or eager-PyTorch fallback.
"""
__version__ = '0.1.0'
