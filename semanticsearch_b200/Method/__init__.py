"""Drop-in counterparts of the reference's ``Method`` package for the dense-similarity path."""
