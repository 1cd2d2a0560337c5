from . import misc, linalg, random      # noqa: F401
