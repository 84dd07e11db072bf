"""Drop-in counterparts of the reference's ``Tool`` package for the dense-similarity path."""
