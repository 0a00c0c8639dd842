"""aarmvsnet_b200: B200-native (sm_100a) AA-RMVSNet per-depth-plane cost-volume pipeline."""
__version__ = "0.1.0"
