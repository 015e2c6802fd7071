from .simulate import NavierStokesSystem  # noqa: F401
