from occnet_b200.plugin.modules import MyCustomBaseTransformerLayer   # noqa: F401
