"""sgg_b200: B200-native (sm_100a) implementation of the Scene-Graph-GAN training hot path."""
__version__ = "0.1.0"
