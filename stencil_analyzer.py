"""Drop-in for the reference's ``stencil_analyzer`` module (own implementation, same variant description)."""
from absinthe_b200.analysis import (analyze_stencil, compute_boundaries, compute_dataflow, count_fetches,  # noqa: F401
                                    parse_stencil, verify_program)
