"""Drop-in pieces of the reference's ``trainer/<method>.py`` plugins that sit on the hot path: the criteria installed
by ``get_criterion`` and the pseudo-label generators.  The training / evaluation loops themselves stay the
reference's (out of scope); a reference ``ActiveTrainer`` picks these up by
inheriting the mixin of the same module name."""
