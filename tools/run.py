"""Entry point with the reference's command shape (``python tools/run.py fit --config <yaml>``), Lightning-free:
see refign_b200/cli.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from refign_b200.cli import main  # noqa: E402

if __name__ == "__main__":
    main(sys.argv[1:])
