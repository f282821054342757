"""nascargymnasium_b200 -- B200-native batched stepping engine for NascarGymnasium's CarEnv.

Only the hot path lives here: track ingestion, the CUDA step
kernels behind the C ABI in ``include/ncg.h``, and the CarEnv / VectorEnv /
VecEnv host mirrors.  There is no CPU fallback: importing the engine without
the built CUDA library raises.
"""
__version__ = "0.1.0"
