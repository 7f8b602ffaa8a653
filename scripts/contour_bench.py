#!/usr/bin/env python
"""Contour-stage benchmark: `python bench.py --leg contours` (CPU only)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.execv(sys.executable, [sys.executable, os.path.join(ROOT, 'bench.py'), '--leg', 'contours'] + sys.argv[1:])
