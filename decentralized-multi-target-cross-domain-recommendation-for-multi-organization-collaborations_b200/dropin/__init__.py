"""Host-side mirror of the reference interface for the hot path: put THIS directory ahead of the reference's
``src`` on ``sys.path`` and the unmodified drivers import ``models``, ``assist`` and ``organization`` from here
(see INTEGRATION.md)."""
