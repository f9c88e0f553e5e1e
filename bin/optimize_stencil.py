#!/usr/bin/env python
"""optimize_stencil.py on the B200 objective (same flags as the reference script)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optimize_stencil_b200.cli import main_optimize
sys.exit(main_optimize())
