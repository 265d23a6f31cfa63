"""blvm_b200 — B200-native (sm_100a) kernels for the DMoL + Gaussian-KL + masked-ELBO path of
JakobHavtorn/benchmarking-lvms, behind the reference's own module / function API.

    import blvm_b200
    lik = blvm_b200.DiscretizedLogisticMixtureDense(x_dim, 1, num_mix=10, num_bins=2**16).cuda()
    params = lik(h)                                    # (B, T, 30) Linear output, carried packed
    out = blvm_b200.fused_elbo(y, params, x_sl, [blvm_b200.KLLevel(mu_q, sd_q, mu_p, sd_p, stride=64)],
                               beta=0.5, free_nats=0.0625, num_bins=lik.num_bins)
    out.loss.backward()

or, with the reference tree importable, `blvm_b200.patch_blvm()` and run `experiments/experiment_*_audio.py` as is.
The CUDA library is mandatory: importing this package without `lib/libblvm_b200.so` raises (no CPU fallback).
"""
from . import _lib  # noqa: F401  (fails loudly if the CUDA library is not built)
from .amp import register_grad_scaler
from .distributed import SumsExchange, all_reduce_sums, bind_to_gpu_numa_node, combine_sums, global_denominator, shard_rows
from .distributions import (ConditionalDistribution, DiagonalGaussianMixtureDense, DiscretizedLogisticDense,
                            DiscretizedLogisticMixtureDense, DLParams, DMoLParams, GMMParams, LinearDMoLParams)
from .elbo import (KLLevel, cwvae_compute_elbo, fused_elbo, pack_dmol_params, srnn_compute_elbo, stcn_compute_loss,
                   vrnn_compute_elbo, wavenet_compute_loss)
from .log_likelihoods import discretized_logistic_ll, discretized_logistic_mixture_ll, gaussian_ll, gaussian_mixture_ll
from .metrics import RunningMean, bits_per_dim, elbo_metrics
from .operations import level_lengths, sequence_mask
from .ops import check_input_range, launch_count, reset_launch_count
from .patch import patch_blvm, unpatch_blvm
from .transforms import Quantize
from .variational import discount_free_nats, kl_divergence_gaussian, kl_divergence_gaussian_mc

__version__ = "0.2.1"
LIB_PATH = _lib.LIB_PATH
