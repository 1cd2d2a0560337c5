"""Namespace mirror of ``bayespy.inference.vmp`` so that ``from bayespy.inference.vmp.transformations import ...``
keeps working after the package rename."""
