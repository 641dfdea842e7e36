/* Empty stand-in for TBB's <scalable_allocator.h>, which the reference's
 * mm/inc/utility.h:11 includes unconditionally.  With -DCPP the reference's
 * my_malloc/my_free (mm/inc/utility.h:126-153) use new[]/delete[] and never
 * touch the TBB API, so nothing needs declaring here. TEST INFRASTRUCTURE ONLY. */
