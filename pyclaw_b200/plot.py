"""``pyclaw.plot`` entry points (src/pyclaw/plot.py).  The reference forwards to the
matplotlib-based visclaw package; plotting is outside the hot-path scope, so these only say so.
Frames are written with ``controller.output_format`` ('ascii' is what visclaw reads)."""


def _no_plot(*args, **kwargs):
    raise NotImplementedError("plotting is outside the scope of pyclaw_b200: write frames with "
                              "controller.output_format = 'ascii' and plot them with visclaw")


plotInteractive = interactive_plot = plotHTML = html_plot = plotPetsc = _no_plot
