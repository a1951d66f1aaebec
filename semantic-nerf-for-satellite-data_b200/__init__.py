"""B200-native (sm_100a) implementation of the SatNeRF / Semantic-NeRF ray-rendering hot path.

Import name: ``semnerf_b200`` (the directory name carries the reference repository's name and is
not a valid Python identifier; ``semnerf_b200/__init__.py`` at the repo root re-exports this package).
"""
__version__ = "0.1.0"
