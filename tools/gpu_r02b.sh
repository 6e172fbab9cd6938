#!/bin/bash
# round 2: the resident drop-in shim against the compiled reference, then the whole GPU suite
timeout 900 python -m pytest tests/test_dropin_resident.py -x -q 2>&1 | tail -25
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
