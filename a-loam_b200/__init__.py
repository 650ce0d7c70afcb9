"""aloam-b200: B200-native A-LOAM per-scan registration hot path."""
