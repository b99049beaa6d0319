class LinearSegmentedColormap:  # placeholder
    pass
