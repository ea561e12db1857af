"""B200-native multi-stream tracking-by-detection hot path (drop-in for BoxMOT's
create_tracker / tracker.update API).  See DESIGN.md."""
__version__ = "0.1.0"

TRACKERS = ["bytetrack", "botsort", "ocsort"]
