"""B200-native implementation of the I3RC Monte Carlo photon-tracing integrator.

The package holds only what the hot path needs: ``csrc/`` (CUDA kernels + the C ABI of
``include/i3rc_b200.h``) and the host-side mirror of the reference's module interface for this path
(module names follow the reference: ErrorMessages, RandomNumbers, scatteringPhaseFunctions,
opticalProperties, surfaceProperties, monteCarloIllumination, monteCarloRadiativeTransfer).
"""
__version__ = "0.1.0"
