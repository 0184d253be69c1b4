def kmeans_sample(*a, **k):  # lightcone jackknife patches only
	raise NotImplementedError("kmeans_radec stand-in: lightcone path is out of scope")
