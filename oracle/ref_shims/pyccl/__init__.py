def __getattr__(name):  # lightcone only
	raise NotImplementedError("pyccl stand-in: lightcone path is out of scope")
