class LambdaCDM:  # imported at module scope by the reference box modules, never used on the box path
	def __init__(self, *a, **k):
		self.args = (a, k)
