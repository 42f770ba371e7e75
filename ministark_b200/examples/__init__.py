"""Example AIRs mirroring the reference's examples/ directory (only the ones BASELINE.json / SURVEY.md §8d use)."""
