"""Fused sweep plans (filled in below)."""


def attach(model):
    return []
