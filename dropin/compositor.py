"""Top-level ``compositor`` module for callers that do ``from compositor import composite``
(macro_placement_test.py:15, agentic/nodes/compositor.py, tests/test_compositor.py:2 of the
reference).  Put this directory before the reference on sys.path to switch the hot path."""
from image_transformation_b200.compositor import composite, load_object_images  # noqa: F401
