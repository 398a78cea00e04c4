"""CPU oracle for the wgpu-mm hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (wgpu_mm_b200) never does.
"""
from .oracle import *  # noqa: F401,F403
