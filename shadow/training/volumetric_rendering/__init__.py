"""Shadow of `training.volumetric_rendering`: the four modules below re-export the B200 drop-ins."""
