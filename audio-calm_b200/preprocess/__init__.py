"""Drop-in mirrors of the reference's ``preprocess/`` modules (same module names, symbols and CLIs)."""
