"""mca_paper_b200 — B200-native (sm_100a) implementation of the mca-paper training hot path."""
