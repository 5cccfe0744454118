"""Drop-in replacement for the reference's `src` package on the TARGET-VAE training hot path (src/models.py)."""
