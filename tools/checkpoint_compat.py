#!/usr/bin/env python3
"""CLI of unidefense_b200.checkpoint:  python tools/checkpoint_compat.py verify best_model.bin --model UDR50 --kwargs '{...}'"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unidefense_b200.checkpoint import main  # noqa: E402

if __name__ == "__main__":
    raise SystemExit(main())
