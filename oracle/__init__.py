"""Test infrastructure: CPU oracle of the SRF routing path. Never imported by srf_b200."""
