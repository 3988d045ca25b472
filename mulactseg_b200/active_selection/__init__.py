"""Drop-in selector plugins: ``importlib.import_module('active_selection.<name>').RegionSelector(args)``
(reference: ``train_AL.py:29-32``).  Module names match the reference one to one."""
