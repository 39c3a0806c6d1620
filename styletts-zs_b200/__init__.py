"""B200-native StyleTTS-ZS inference hot path: style-diffusion sampling + duration predictor.

Public API: ``StyleTTSZSPath.sample_style`` / ``predict_duration`` backed by the
C-ABI library ``csrc/libstz.so`` (hand-written sm_100a kernels).  There is no CPU fallback: the
constructor raises if the library or an sm_100 device is missing.
"""
from .spec import (StzConfig, DEFAULT, TINY, init_weights, view_weights, weight_offsets, weight_entries,
                   synthetic_inputs, n_noise_slices, SAMPLER_STUDENT, SAMPLER_TEACHER, SAMPLER_GUIDED, ABI_VERSION)
from .path import StyleTTSZSPath, load_library, StzError, philox_normal, sampler_plan  # noqa: F401
from .shard import (shard_utterances, take_shard, synthesize_sharded, synthesize_sharded_shm,  # noqa: F401
                    SharedHostOutputs)
