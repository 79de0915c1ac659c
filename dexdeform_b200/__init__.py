"""dexdeform_b200 -- Blackwell-native differentiable MLS-MPM substep + adjoint behind DexDeform's operator boundary."""
__version__ = "0.1.0"
