"""KITTI-MV3D dataset feed (lib/datasets of the reference): the step immediately before the hot path."""
import os

ROOT_DIR = os.environ.get("MV3D_ROOT_DIR", os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
