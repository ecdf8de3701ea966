"""duoformer_tcga_b200 — B200-native (sm_100a) DuoFormer multi-scale transformer forward path."""
__version__ = "0.1.0"
