"""B200-native TGN training / scoring hot path of PfoTGNRec."""
__version__ = "0.1.0"
