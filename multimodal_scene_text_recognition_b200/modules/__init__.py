"""nn.Module mirrors of the reference's ``modules/`` package (hot-path classes only)."""
