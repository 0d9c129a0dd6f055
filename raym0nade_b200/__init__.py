"""raym0nade_b200 - B200-native (sm_100a CUDA) implementation of Raym0nade's path-tracing hot path."""
