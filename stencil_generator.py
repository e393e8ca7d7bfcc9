"""Drop-in for the reference's ``stencil_generator`` module: same names, sm_100a code generation."""
from absinthe_b200.generator import (build_experiment, compute_schedule, generate_code, generate_makefile,  # noqa: F401
                                     generate_script, parse_results, write_results)
