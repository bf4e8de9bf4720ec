"""Mirror of the reference's ``code/networks`` package (``encoder3d``, ``headnerf``)."""
