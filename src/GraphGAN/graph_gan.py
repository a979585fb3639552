"""Drop-in for the reference's src/GraphGAN/graph_gan.py: ``python graph_gan.py`` from src/GraphGAN."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from graphgan_b200.graph_gan import GraphGAN  # noqa: E402,F401

if __name__ == "__main__":
    graph_gan = GraphGAN()
    graph_gan.train()
