sentinel = object()
