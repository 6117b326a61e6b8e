from .render import *  # noqa: F401,F403  (same star-export as the reference's src/render/__init__.py)
from .render import render, render_chunk, run_network, raw2outputs, sample_pdf, compute_tv_regularization, sample_points, sample_fine  # noqa: F401
