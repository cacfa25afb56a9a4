"""rlcf_b200 -- B200-native (sm_100a) implementation of RLCF's per-sample test-time-adaptation hot path."""
__version__ = "0.1.0"
