python scripts/probe/oct_cycles.py
