"""Top-level ``background_resizing`` module for callers that do
``from background_resizing import fill_solid`` (macro_placement_test.py:14,
agentic/nodes/compositor.py:11 of the reference)."""
from image_transformation_b200.background_resizing import (  # noqa: F401
    _axis_variance,
    _edge_strip_median_colors,
    _load_background_rgba,
    _median_color_nontransparent,
    fill_gradient,
    fill_solid,
)
