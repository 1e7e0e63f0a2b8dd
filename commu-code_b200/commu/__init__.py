"""commu: B200-native drop-in for the ComMU Transformer-XL hot path (train / generate)."""
