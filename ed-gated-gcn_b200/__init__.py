"""B200-native gated-GCN hot path."""
