"""gblastn_b200 — B200-native blastn preliminary-search hot path (scan → extend → gapped)."""
__version__ = "0.1.0"
