"""pyclaw_b200 -- PyClaw's finite-volume time-step hot path, native on B200 (sm_100a)."""
