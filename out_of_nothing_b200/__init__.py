"""out_of_nothing_b200 — B200-native world update for Out_of_Nothing (hot path only)."""
