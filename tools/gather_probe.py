#!/usr/bin/env python
"""Runs tools/libgather_probe.so (built on the authoring box from tools/gather_probe.cu:
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC tools/gather_probe.cu -o tools/libgather_probe.so)."""
import ctypes
import os
import sys

lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgather_probe.so"))
sys.stdout.flush()
lib.gather_probe_main()
