class ProcessingPool:
	"""Serial stand-in for pathos' pool (the reference's light-cone jackknife maps its brute loops over patches with it,
	measure_jackknife.py:381-431): same results, one process."""

	def __init__(self, nodes=1, *a, **k):
		self.nodes = nodes

	def map(self, func, *iterables):
		return [func(*args) for args in zip(*iterables)]
