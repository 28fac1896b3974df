"""Tensor / sequence / weight parallel building blocks and the peer-memory (NVLink) back-ends."""


def reset_caches() -> None:
    """Forget every process-group-keyed cache of the peer-memory back-ends (called by ``ParallelContext.destroy``)."""
    import sys

    mods = sys.modules
    if "internevo_b200.parallel.fused" in mods:
        mods["internevo_b200.parallel.fused"]._tp_backends.clear()
        mods["internevo_b200.parallel.fused"]._isp_backends.clear()
    if "internevo_b200.parallel.linear" in mods:
        mods["internevo_b200.parallel.linear"].set_fused_backend(None)
    if "internevo_b200.parallel.moe_fused" in mods:
        mods["internevo_b200.parallel.moe_fused"]._backends.clear()
    if "internevo_b200.parallel.sp_attention" in mods:
        mods["internevo_b200.parallel.sp_attention"].reset()
    if "internevo_b200.parallel.symm" in mods:
        mods["internevo_b200.parallel.symm"]._flags_cache.clear()
        mods["internevo_b200.parallel.symm"]._intra_cache.clear()
        mods["internevo_b200.parallel.symm"]._node_ids = None
