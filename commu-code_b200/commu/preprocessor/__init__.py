"""Only the token-vocabulary constants of the reference's preprocessor are needed by the hot
path's callers; MIDI parsing / encoding is out of scope."""
