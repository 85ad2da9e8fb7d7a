"""B200-native host of the quantized Wan2.1 DiT hot path (block runtime, sequence parallelism, calibration)."""
