"""Vocabulary layout constants (facts of the ComMU token format; reference
commu/preprocessor/encoder/event_tokens.py:308-329).  Vocabulary size 729, pad/BOS 0, EOS 1."""
import enum

_LAYOUT = dict(EOS=1, BAR=2, PITCH=3, NOTE_VELOCITY=131, CHORD_START=195, CHORD_END=303,
               NOTE_DURATION=304, POSITION=432, BPM=560, KEY=601, TS=626, PITCH_RANGE=630,
               NUM_MEASURES=638, INST=641, GENRE=650, VELOCITY=653, TRACK_ROLE=719, RHYTHM=726,
               REMI_META_OFFSET=138, META_CC_OFFSET=7, VOCAB_SIZE=729)
TOKEN_OFFSET = enum.Enum("TOKEN_OFFSET", _LAYOUT)
