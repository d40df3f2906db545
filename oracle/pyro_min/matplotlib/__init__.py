"""Import stub: reference models/ivae.py:15 and utils/viz.py:4 import
matplotlib.pyplot, which is absent in this image.  Plotting is out of scope."""
