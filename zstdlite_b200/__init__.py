"""zstdlite_b200: B200-native Zstandard codec behind zstdlite's API."""
