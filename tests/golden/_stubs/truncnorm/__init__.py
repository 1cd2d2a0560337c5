"""Stub so the read-only reference imports here (truncnorm is only used for truncated Gaussians)."""
def moments(*a, **k):
    raise RuntimeError("truncnorm stub")
