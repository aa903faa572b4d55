python scripts/probe/stage_times_small.py
python scripts/probe/latency_trace.py
