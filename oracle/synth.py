"""Synthetic workload generator: moved to the repository root (`synth_workload.py`) so that the product's bench arm and tools do
not import anything from `oracle/` (which is test infrastructure).  Re-exported here for the oracle's own scripts and the tests."""
from synth_workload import *          # noqa: F401,F403
from synth_workload import state_dict_spec, make_state_dict, make_photo, make_doc_inputs, make_map64      # noqa: F401
