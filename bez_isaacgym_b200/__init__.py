"""B200-native BezKick hot path."""
