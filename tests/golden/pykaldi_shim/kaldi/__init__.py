"""numpy stand-in for the few pykaldi entry points shennong's plp.py uses (golden generation only)."""
