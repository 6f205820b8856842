"""coclr_b200 -- sm_100a kernels + host executor behind the CoCLR pre-training hot path."""
__version__ = "0.1.0"
