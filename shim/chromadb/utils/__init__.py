from . import embedding_functions  # noqa: F401
