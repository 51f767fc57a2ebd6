"""Minimal attribute-access dict: stand-in for the ``easydict.EasyDict`` the reference's config
factories return (pcl_segmentation/configs/*.py; ``easydict`` is not installed in this image)."""


class EasyDict(dict):
  def __getattr__(self, name):
    try:
      return self[name]
    except KeyError as e:
      raise AttributeError(name) from e

  def __setattr__(self, name, value):
    self[name] = value

  def __delattr__(self, name):
    del self[name]
