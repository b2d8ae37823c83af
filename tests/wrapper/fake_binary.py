#!/usr/bin/env python
"""TEST INFRASTRUCTURE: the CmdStan shim with the device classes replaced by the oracle-backed stand-ins of
tests/test_vinit_host.py, so that the reference's own bash wrappers can be run end to end on a box without a GPU.
Invoked through a symlink named like the model (the basename selects the family, as for the real shim)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.realpath(__file__))))
sys.path.insert(0, ROOT)
from csplotch_b200 import api, cli
from tests import test_vinit_host as fakes

api.Model, api.Batch, api.Sampler = fakes._Model, fakes._Batch, fakes._Sampler
api.check = lambda rc: None
sys.exit(cli.cmdstan_shim(sys.argv))
