Gpc = sr = None
