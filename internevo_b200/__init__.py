"""internevo_b200: a Blackwell (sm_100a) native hybrid-parallel LLM pre-training engine with InternEvo's capabilities."""
__version__ = "0.1.0"
