"""TEST INFRASTRUCTURE -- the deterministic synthetic weights / images used by the parity tests.
The generator itself (integer hash -> exact float32, no model arithmetic) lives in
`resdsic_b200/utils/synthetic.py`, because the benchmark's own arm draws its synthetic workload
from it and product-side code must not import `oracle/`; this module re-exports it unchanged."""
from resdsic_b200.utils.synthetic import *  # noqa: F401,F403
from resdsic_b200.utils.synthetic import (hash_symmetric, hash_uniform, make_image, make_state_dict,  # noqa: F401
                                          relative_position_index, scale_table, state_dict_spec, synth_state_dict)
