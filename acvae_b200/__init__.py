"""acvae_b200 -- B200-native AC-VAE latent word-decoding step.

The public surface mirrors the reference's `models` package for this path
(`Hybrid_VAEModel`, `VAEModel`, `VAERNNBahdanauAttnDecoder`, `PosteriorRNN_hybrid`,
`PosteriorRNN`, `PriorRNN`, `Seq2SeqAttention`, `CaptionModel`) plus the two loss
callables of the runner boundary.  Importing this package does not need a GPU;
running it does, and needs `libacvae_b200.so` (there is no fallback).
"""
import os as _os

# The train step forks up to ten concurrent streams (critical chain, posterior directions, prior, memory backward,
# weight-gradient fan).  With the default of 8 hardware work queues several of them alias onto one queue and a
# critical kernel can sit behind an unrelated batched GEMM (seen as 30 us gaps in profiles/r1 timelines).  Must be
# set before the CUDA context exists, hence at import; an explicit user setting wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import synthetic  # noqa: F401
from .models import (CaptionModel, Hybrid_VAEModel, PosteriorRNN, PosteriorRNN_hybrid, PrecomputedEncoder,  # noqa: F401
                     PreparedBatch, PriorRNN, Seq2SeqAttention, VAEModel, VAERNNBahdanauAttnDecoder)
from .train_util import CrossEntropyLoss, FusedVAELoss, LabelSmoothingLoss, Normal_kl_loss  # noqa: F401
from .lazy import LazyLogits  # noqa: F401
from .optim import DistributedClipAdam, FusedClipAdam, PeerMemoryUnavailable  # noqa: F401
from .metrics import diversity_stats, ids_to_sentences, mbleu, predictions_json  # noqa: F401
from .sampler import GraphSampler, gather_captions  # noqa: F401
from .functional import encoder_handoff, get_precision, set_input_event, set_precision  # noqa: F401

# the reference resolves decoders as getattr(models.decoder, name) and posteriors/priors as
# getattr(text_encoder, name) (pytorch_runner_vae.py:44, vae_model.py:678-691): same attribute paths
from . import models as decoder  # noqa: F401,E402
from . import models as text_encoder  # noqa: F401,E402
