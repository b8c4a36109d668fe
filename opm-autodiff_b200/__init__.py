"""opm-autodiff_b200: B200-native ILU0-BiCGSTAB backend for OPM Flow's linear solve (hot path only)."""
