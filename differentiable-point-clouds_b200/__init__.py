"""B200-native differentiable point-cloud projection path.
Import as `dpc_b200` (alias package; a directory name with a hyphen is not importable)."""
