"""B200-native bundle-adjustment hot path for fiducial-marker camera calibration."""
__version__ = "0.1.0"
