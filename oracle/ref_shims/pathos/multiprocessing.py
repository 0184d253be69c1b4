class ProcessingPool:  # lightcone jackknife only; not on the periodic-box path
	def __init__(self, *a, **k):
		raise NotImplementedError("pathos stand-in: lightcone path is out of scope")
