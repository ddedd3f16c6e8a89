"""The motion-search parity cases live in tools/me_cases.py (bench.py's secondary measurement uses them too)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from me_cases import *  # noqa: E402,F401,F403
