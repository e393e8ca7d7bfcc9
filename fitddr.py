#!/usr/bin/env python
"""fitddr experiment driver on the B200 backend -- same flags as the reference's fitddr.py
(see absinthe_b200/driver.py)."""
import sys

from absinthe_b200.driver import main

if __name__ == "__main__":
    main("fitddr", sys.argv[1:])
