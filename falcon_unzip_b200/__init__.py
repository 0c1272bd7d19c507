"""B200-native read-phasing hot path of FALCON-Unzip."""
__version__ = "0.1.0"
