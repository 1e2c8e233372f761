"""jamie_b200: B200-native implementation of JAMIE's coupled-VAE training step and inference path."""
__version__ = '0.1.0'
