"""Empty stand-in: the reference imports transforms3d at module scope but the hot path never calls it."""
euler = None
