"""B200-native (sm_100a) implementation of style_transfer_visualizer's optimisation hot path."""

__version__ = "0.1.0"
