"""B200-native `bootstrapper/post` hot path (ws): same module / function names as the reference."""
