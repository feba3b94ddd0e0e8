"""Mirror of the reference's ``common/nets`` package for the modules on the hot path."""
