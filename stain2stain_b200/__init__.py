"""B200-native hot path of nirschl-lab/stain2stain."""


def invalidate_caches():
    """Drop every cached derived object: packed 16-bit GEMM operands and captured sampler graphs.

    Staleness of both is detected through `tensor._version`, which in-place ops (`copy_`, `add_`, optimizers,
    `load_state_dict`) bump -- but writes through `p.data` (EMA weight swaps, legacy optimizers, manual `p.data.copy_`) do NOT.
    Code that updates parameters that way must call this afterwards, otherwise the engine keeps computing with the old
    weights while `state_dict()` shows the new ones."""
    from . import neural_ode, ops
    ops.PACK_CACHE.clear()
    neural_ode.clear_graphs()
