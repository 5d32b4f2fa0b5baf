"""Runs bench.guided_matchers once (for ncu launch lists of the small matchers)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import openmavis_b200 as ob
print(json.dumps(bench.guided_matchers(ob, 0)))
