"""example_block.py of the reference (same model in gtn.block format): examples/example.py with --block."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from example import main  # noqa: E402

if __name__ == "__main__":
    main(force_block=True)
