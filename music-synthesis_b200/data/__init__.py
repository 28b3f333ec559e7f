from .datastore import DeviceAudioStore, batch_stream, draw_crops  # noqa: F401
