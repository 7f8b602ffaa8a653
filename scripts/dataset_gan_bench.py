#!/usr/bin/env python
"""DatasetGAN labeller benchmark: `python bench.py --leg dataset_gan` (needs a B200)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.execv(sys.executable, [sys.executable, os.path.join(ROOT, 'bench.py'), '--leg', 'dataset_gan'] + sys.argv[1:])
