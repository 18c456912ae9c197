"""``python demo.py {manager|worker} <host:port> <port>`` -- reference-compatible CLI."""
import sys

from baton_b200.demo import LinearTestWorker, main, make_app  # noqa: F401
from baton_b200.models import LinearModel as Model  # noqa: F401  (reference name)

if __name__ == "__main__":
    main(sys.argv[1:])
