"""B200-native probabilistic semantic mapping (project -> label lookup -> Bayesian BEV update -> render)."""
__version__ = "0.1.0"
