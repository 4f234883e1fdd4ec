"""hupr_b200 — B200-native (sm_100a) implementation of the HuPR radar->pose hot path.

The directory is named after the upstream project; import it as ``hupr_b200`` (the tiny alias package at
the repo root maps that name onto this directory).
"""
__version__ = "0.1.0"
