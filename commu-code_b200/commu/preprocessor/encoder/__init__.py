from .event_tokens import TOKEN_OFFSET  # noqa: F401
