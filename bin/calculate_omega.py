#!/usr/bin/env python
"""calculate_omega.py on the B200 objective (same flags as the reference script)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimize_stencil_b200.cli import main_calculate
sys.exit(main_calculate())
