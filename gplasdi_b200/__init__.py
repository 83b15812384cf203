"""gplasdi_b200 -- B200-native (sm_100a) greedy-sampling and SINDy-calibration hot path of
LLNL/GPLaSDI behind the reference's own Python plugin API.

Registries use the reference's YAML `type` keys (reference src/lasdi/workflow.py:15-21) so that a
workflow can select these implementations the same way:
    trainer_dict['gplasdi'], latent_dict['ae'], ld_dict['sindy']
Importing the package does not touch the GPU; the CUDA library (gplasdi_b200/lib/
libglasdi_b200.so, C ABI in include/glasdi_b200.h) is loaded on first use and there is no CPU
fallback.
"""
__version__ = "0.1.0"


def __getattr__(name):
    # lazy: keeps `import gplasdi_b200` light and free of torch/sklearn import cost
    if name in ("trainer_dict", "latent_dict", "ld_dict"):
        from .gplasdi import BayesianGLaSDI
        from .latent_space import Autoencoder
        from .latent_dynamics.sindy import SINDy
        return {"trainer_dict": {"gplasdi": BayesianGLaSDI}, "latent_dict": {"ae": Autoencoder},
                "ld_dict": {"sindy": SINDy}}[name]
    raise AttributeError(name)
