"""Only referenced in type annotations of wosac_post_processing.py (evaluated at class-definition time)."""


def __getattr__(name):  # any message type resolves to a placeholder class
    return type(name, (), {})
